#!/usr/bin/env bash
# "Installs" the UNMODIFIED reference into the git-ignored baseline/_ref/ so that it travels to the GPU box with the
# gpurun snapshot (the tree is pure Python without setup.py / pyproject.toml, so `pip install --target baseline/_ref
# /root/reference` has nothing to build: the install is a verbatim copy of the source tree, byte for byte).
# Used by: tests/test_gpu_reference.py (the reference's own L1-L5 code driving the native plugins on the B200) and
# `bench.py --impl reference` (kind "reference").  Nothing under vsiquantization_b200/ reads it.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${VSIQ_REFERENCE_SRC:-/root/reference}"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/quantizers" ]; then
    echo "install_reference: $SRC not present (GPU box?): keeping $DST as it is"
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
(cd "$SRC" && tar --exclude=.git --exclude=__pycache__ -cf - .) | (cd "$DST" && tar -xf -)
(cd "$SRC" && find . -type f -not -path './.git/*' -not -name '*.pyc' | sort | xargs sha256sum) > "$DST/.SHA256SUMS"
echo "installed reference ($(find "$DST" -type f -name '*.py' | wc -l) python files) into $DST"
