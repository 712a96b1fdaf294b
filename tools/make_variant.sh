#!/usr/bin/env bash
# Build a patched copy of the kernels as tools/_bin/libd3f_<name>.so for A/B measurements on the GPU box:
#
#   tools/make_variant.sh <name> <patch.py>     # patch.py receives the scratch tree's path as argv[1] and edits it
#   D3F_LIBRARY=$PWD/tools/_bin/libd3f_<name>.so python tools/bench_matrix.py --only cfg2a_grid,cfg2b ...
#   D3F_LIBRARY=...                               python -m pytest tests/test_parity_gpu.py -q -k "golden or V4_C1024"
#
# Several variants can be built here (no GPU needed) and compared in ONE gpurun call; tools/_bin/ is git-ignored but
# travels with the working tree.  This is how every row of DESIGN.md §4.7 was measured.
set -euo pipefail
name=$1; patch=${2:-}
root=$(cd "$(dirname "$0")/.." && pwd)
work=$(mktemp -d /tmp/d3f_variant_${name}_XXXX)
mkdir -p "$work/d3fields_b200" "$root/tools/_bin"
cp -r "$root/d3fields_b200/csrc" "$work/d3fields_b200/"
cp -r "$root/include" "$work/"
if [ -n "$patch" ]; then python "$patch" "$work"; fi
cd "$work"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas=-v -shared -Xcompiler -fPIC -I include \
     -o "$root/tools/_bin/libd3f_${name}.so" d3fields_b200/csrc/d3f_abi.cu > build.log 2>&1 || { tail -30 build.log; exit 1; }
grep -A3 "field_tile_kernelILb0ELi0ELb1" build.log | grep -E "registers|spill" || true
echo "built $root/tools/_bin/libd3f_${name}.so (scratch tree: $work)"
