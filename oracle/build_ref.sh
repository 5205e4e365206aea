#!/usr/bin/env bash
# Test infrastructure only (never imported by the product path).
# Builds the UNMODIFIED reference (fluidgym 0.1.2: python package + its PISOtorch CUDA
# extension) for sm_100 into baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).
# Sources are compiled from a scratch copy of /root/reference (the reference tree is read-only
# and setuptools writes build/ next to setup.py); nothing from the reference enters git history.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF_SRC:-/root/reference}"
OUT="$REPO/baseline/_ref"
if [ ! -d "$REF" ]; then echo "reference tree $REF absent: keeping prebuilt $OUT"; exit 0; fi
if ls "$OUT"/fluidgym/simulation/extensions/PISOtorch*.so >/dev/null 2>&1 && [ -z "${FORCE:-}" ]; then
  echo "baseline/_ref already built"; exit 0; fi
TMP="$(mktemp -d /tmp/fluidgym_ref_XXXX)"
cp -r "$REF"/. "$TMP"/
rm -rf "$TMP/.git"
export TORCH_CUDA_ARCH_LIST="10.0" MAX_JOBS="${MAX_JOBS:-6}" CUDA_HOME=/usr/local/cuda
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$OUT" --upgrade "$TMP" 2>&1 | tail -n 40
rm -rf "$TMP"
ls -la "$OUT"/fluidgym/simulation/extensions/*.so
