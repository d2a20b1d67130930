#!/usr/bin/env bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored) for `bench.py --impl reference`.
# 1) the documented offline pip install; the reference's setup.py has no package list and
#    setuptools refuses its flat layout ("Multiple top-level packages discovered"), so
# 2) fall back to what `pip install -e .` would have provided — the source tree itself on
#    sys.path — by copying the python packages verbatim (notebooks/plots/trained artefacts skipped).
set -u
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${REFERENCE_SRC:-/root/reference}"
DST="$HERE/_ref"
[ -d "$DST/optimizers" ] && { echo "reference already installed at $DST"; exit 0; }
[ -d "$SRC" ] || { echo "reference source $SRC not found"; exit 1; }
TMP="$(mktemp -d)"; cp -r "$SRC" "$TMP/src"; chmod -R u+w "$TMP/src"
if python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target "$DST" "$TMP/src" >"$TMP/pip.log" 2>&1 \
   && [ -d "$DST/optimizers" ]; then
  echo "pip install ok"
else
  echo "pip install failed ($(grep -m1 -o 'Multiple top-level packages[^.]*' "$TMP/pip.log" || echo see log)); copying source tree"
  rm -rf "$DST"; mkdir -p "$DST"
  for d in models optimizers problems utils experiments floorplans; do cp -r "$SRC/$d" "$DST/"; done
  mkdir -p "$DST/RL"; cp -r "$SRC/RL/dist_rl" "$SRC/RL/pettingzoo" "$SRC/RL"/*.py "$DST/RL/" 2>/dev/null
  rm -rf "$DST/RL/dist_rl/trained" "$DST/RL/dist_rl/vids"
  cp "$SRC/__init__.py" "$SRC/setup.py" "$SRC/LICENSE" "$DST/" 2>/dev/null
  find "$DST" -name '*.ipynb' -delete
fi
rm -rf "$TMP"
echo "reference at $DST"
