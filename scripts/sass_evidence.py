"""Static SASS mnemonic counts per kernel of the in-tree extension -> profiles/sass_evidence.md (runs on the CPU box)."""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = glob.glob(os.path.join(ROOT, "nn_distributed_training_b200", "ops", "_C*.so"))[0]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UBLKCP", "SYNCS", "LDGMC", "HMMA", "UCGABAR", "LDGSTS", "ATOMS", "RED", "FFMA", "DFMA",
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

# ---- real SASS excerpts (not just counts): every tensor-core / TMEM / TMA / mbarrier / cluster-barrier / multimem instruction of
#      the Blackwell-native kernels with its address, plus the complete first tcgen05 issue loop --------------------------------
PICK = re.compile(r"UTC[A-Z]*MMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UTMALDG|UBLKCP|SYNCS|UCGABAR|LDGMC|MULTIMEM|ELECT|R2UR|CCTL")
WANT = {"mnist_tc_train_kernel<32>": "sass_mnist_tc.txt", "mlp_train_kernel<256, 2>": "sass_mlp_train.txt",
        "dinno_update_kernel<float>": "sass_dinno_update.txt", "mnist_cl64_train_kernel<32>": "sass_mnist_cl64.txt"}
cur, buf = None, []
blocks = {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            blocks[cur] = buf
        cur, buf = m.group(1), []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        buf.append(line.rstrip())
if cur:
    blocks[cur] = buf
names = list(blocks)
dem2 = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
for mangled, lines in blocks.items():
    d = dem2[mangled].replace("nndt::", "")
    for key, fn in WANT.items():
        if key in d:
            sel = [l for l in lines if PICK.search(l)]
            first = next((i for i, l in enumerate(lines) if re.search(r"UTC[A-Z]*MMA", l)), None)
            head = [f"// {d}", f"// cuobjdump -sass of the in-tree extension (sm_100a); {len(lines)} instructions in total.",
                    "// Part 1: every tensor-core (UTC*MMA = tcgen05.mma), TMEM (LDTM = tcgen05.ld, UTCATOMSWS = alloc), TMA (UTMALDG = "
                    "cp.async.bulk.tensor, UBLKCP = cp.async.bulk), mbarrier (SYNCS), cluster barrier (UCGABAR), multimem (LDGMC) instruction.", ""]
            body = head + sel
            if first is not None:
                body += ["", "// Part 2: the instruction stream around the first tcgen05.mma (descriptor set-up in uniform registers, issue, commit)", ""]
                body += lines[max(0, first - 40): first + 60]
            open(os.path.join(ROOT, "profiles", fn), "w").write("\n".join(body) + "\n")
            print("wrote", fn, len(sel), "selected lines")
print("\n".join(out[6:]))
