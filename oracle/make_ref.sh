#!/bin/bash
# TEST INFRASTRUCTURE.  Stage the UNMODIFIED reference (pure Python: package + its own unittest files) under the
# git-ignored oracle/_ref/ so that it travels to the GPU box with the gpurun snapshot (/root/reference does not exist
# there).  Nothing is edited; no reference source enters git history (oracle/_ref/ is in .gitignore).
# Used by: tests/test_reference_conformance.py (the reference's 31 tests run against pocomc_b200; the reference's own
# Sampler with pocomc_b200's Flow and MCMC kernels patched in) and bench.py --impl reference / cpu_baseline
# (kind "reference": the reference's mcmc.py over oracle/zuko, its only missing dependency).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/pocomc" ]; then
  echo "make_ref.sh: $SRC/pocomc not found (fine on the GPU box: oracle/_ref ships prebuilt)"; exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC/pocomc" "$HERE/_ref/pocomc"
cp -r "$SRC/tests" "$HERE/_ref/tests"
find "$HERE/_ref" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$HERE/_ref/COMMIT"
echo "staged $(find "$HERE/_ref" -name '*.py' | wc -l) reference files under oracle/_ref"
