#!/usr/bin/env bash
# Pull result directories from a training host (reference: rsync_results.sh).
# usage: ./rsync_results.sh user@host:/path/to/repo/results [local_dir]
set -euo pipefail
SRC="${1:?remote results path, e.g. user@gpu-box:/work/repo/results}"
DST="${2:-./results}"
mkdir -p "$DST"
rsync -avz --include='*/' --include='*.pt' --include='*.yaml' --include='*.gpickle' --include='*.npy' --include='*.npz' \
      --exclude='*' "$SRC/" "$DST/"
