#!/bin/sh
# Snapshot of the UNMODIFIED Python reference for the benchmark's reference arm and the parity tests.
#
#   sh oracle/make_ref.sh          # /root/reference/plenopticam -> oracle/_ref/plenopticam
#
# oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history) but NOT
# gpurun-ignored, so the snapshot travels to the GPU box with the built .so files; /root/reference itself does not
# exist there.  The files are byte-for-byte copies (checked below); everything the reference needs that this image
# lacks (tkinter, imageio, scikit-image ..., the removed scipy interp2d) is supplied at import time by
# oracle/ref_shim.py, not by editing the copy.  TEST / BENCHMARK INFRASTRUCTURE ONLY.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${PLENOPTICAM_REFERENCE_SRC:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/plenopticam" ]; then
    echo "make_ref: no reference at $SRC (nothing to do; an existing $DST is kept)" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
# the package only: no docs, notebooks, tests or binary blobs
(cd "$SRC" && find plenopticam \( -name '*.py' -o -name '*.json' \) -not -path '*__pycache__*' -exec cp --parents {} "$DST" \;)
cp "$SRC/LICENSE.rst" "$DST/LICENSE.rst"
(cd "$SRC" && find plenopticam -name '*.py' | sort | xargs sha256sum) > "$DST/SHA256SUMS.src"
(cd "$DST" && find plenopticam -name '*.py' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
cmp "$DST/SHA256SUMS.src" "$DST/SHA256SUMS"
rm "$DST/SHA256SUMS.src"
echo "make_ref: $(wc -l < "$DST/SHA256SUMS") python files copied unmodified to $DST"
