#!/bin/sh
# oracle/_ref: a byte-identical copy of the reference's own Python sources and hyperparameter YAMLs, taken from
# where they lie under /root/reference.  The reference is pure Python, so "building" it is copying it; nothing is
# edited.  oracle/_ref is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored:
# it travels to the GPU box like the built .so, where bench.py --impl reference and the reference-backed tests
# import it through oracle/ref_shim.py.  oracle/ref_manifest.sha256 (tracked) pins the content.
#
#   sh oracle/make_ref.sh [/root/reference]
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
[ -d "$SRC/rl_algo_impls" ] || { echo "no reference at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
(cd "$SRC" && find rl_algo_impls \( -name '*.py' -o -name '*.yml' \) -type f -print0 | tar --null -cf - -T -) | (cd "$DST" && tar -xf -)
(cd "$DST" && find rl_algo_impls -type f | LC_ALL=C sort | xargs sha256sum) > "$DST/MANIFEST.sha256"
if [ "$2" = "--pin" ] || [ ! -f "$HERE/ref_manifest.sha256" ]; then cp "$DST/MANIFEST.sha256" "$HERE/ref_manifest.sha256"; fi
cmp -s "$DST/MANIFEST.sha256" "$HERE/ref_manifest.sha256" || { echo "oracle/_ref differs from oracle/ref_manifest.sha256" >&2; exit 2; }
echo "oracle/_ref: $(wc -l < "$DST/MANIFEST.sha256") files, $(du -sh "$DST" | cut -f1)"
