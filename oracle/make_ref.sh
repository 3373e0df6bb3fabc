#!/usr/bin/env bash
# oracle/make_ref.sh — TEST INFRASTRUCTURE: vendors the UNMODIFIED Python reference into
# oracle/_ref/ so that it travels to the GPU box (oracle/_ref/ is git-ignored, not
# gpurun-ignored: nothing of the reference enters this repository's history).
#
# The reference (JonasBreuling/scipy_dae v0.1.0) is pure Python; the hot path is five modules
# (scipy_dae/integrate/_dae/{dae,base,radau,bdf,common}.py) plus three __init__.py files.  They are
# copied byte for byte from the read-only tree, and SHA256SUMS records what was copied so that
# oracle/reference.py can prove at run time that the files it imports are the untouched ones.
#
# Used by (and only by): bench.py's cpu_baseline / --impl reference legs, tests/ (live-reference
# parity on the GPU box, the DaeSolver adapter through the reference's own solve_dae) and
# __graft_entry__.build().  scipy_dae_b200/ never imports it.
#
#   usage: oracle/make_ref.sh [/path/to/reference]      (default /root/reference)
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
FILES=(
  scipy_dae/__init__.py
  scipy_dae/integrate/__init__.py
  scipy_dae/integrate/_dae/__init__.py
  scipy_dae/integrate/_dae/dae.py
  scipy_dae/integrate/_dae/base.py
  scipy_dae/integrate/_dae/radau.py
  scipy_dae/integrate/_dae/bdf.py
  scipy_dae/integrate/_dae/common.py
)
if [ ! -d "$SRC/scipy_dae" ]; then
  echo "make_ref.sh: $SRC/scipy_dae not found (the reference tree only exists in the build container)" >&2
  exit 3
fi
rm -rf "$DST"
for f in "${FILES[@]}"; do
  mkdir -p "$DST/$(dirname "$f")"
  cp "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum "${FILES[@]}" ) > "$DST/SHA256SUMS"
( cd "$DST" && sha256sum --quiet -c SHA256SUMS )
echo "oracle/_ref: ${#FILES[@]} files of the unmodified reference, checksums in oracle/_ref/SHA256SUMS"
