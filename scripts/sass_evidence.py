"""Static SASS mnemonic counts per kernel of the in-tree extension -> profiles/sass_evidence.md (runs on the CPU box)."""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = glob.glob(os.path.join(ROOT, "nn_distributed_training_b200", "ops", "_C*.so"))[0]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "LDGMC", "HMMA", "UCGABAR", "LDGSTS", "ATOMS", "RED", "FFMA", "DFMA",
        "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "BAR"]
rows, cur, cnt = [], None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        cnt["UCGABAR" if op.startswith("UCGABAR") else op] += 1
        cnt["total"] += 1
if cur:
    rows.append((cur, cnt))
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
out = ["# SASS evidence (cuobjdump -sass of the in-tree extension, sm_100a)", "",
       "Static instruction counts per kernel (`python scripts/sass_evidence.py`).  `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit,",
       "`UTCATOMSWS` = TMEM alloc/dealloc, `UBLKCP` = cp.async.bulk (TMA engine), `SYNCS` = mbarrier ops, `LDGMC` = multimem.ld_reduce (NVLS in-switch",
       "reduction), `HMMA` = mma.sync (the 3xTF32 fc1 / da1 / dW1 GEMMs of the MNIST kernels), `UCGABAR` = barrier.cluster (one-launch round kernel).", "",
       "| kernel | " + " | ".join(COLS) + " | total |", "|---|" + "---|" * (len(COLS) + 1)]
for (name, c), d in sorted(zip(rows, dem), key=lambda t: t[1]):
    if not any(k in d for k in ("nndt::",)):
        continue
    d = d.replace("nndt::", "")
    out.append(f"| `{d[:100]}` | " + " | ".join(str(c.get(k, 0)) for k in COLS) + f" | {c['total']} |")
open(os.path.join(ROOT, "profiles", "sass_evidence.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[6:]))
