#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's own CUDA kernels (read-only, where they lie under $REFERENCE) for
# sm_100a into oracle/_ref/libref_kernels.so — a secondary on-GPU comparator (fp16-accumulating, bs=1 scratch).
# ~5 minutes (torch/extension.h is pulled in by the reference's core/Scalar.cuh).  Outputs only into oracle/_ref/.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REFERENCE="${REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
PY=${PYTHON:-python}
TORCH_INC=$($PY - <<'PY'
from torch.utils.cpp_extension import include_paths
print(" ".join("-I" + p for p in include_paths()))
PY
)
TORCH_LIB=$($PY -c "import torch, os; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))")
PY_INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -ftz=true -std=c++17 --expt-relaxed-constexpr -shared -Xcompiler -fPIC \
  -I "$REFERENCE/scripts/modeldb/bindings" $TORCH_INC -I "$PY_INC" -D_GLIBCXX_USE_CXX11_ABI=1 \
  -L "$TORCH_LIB" -lc10 -ltorch_cpu -Xlinker -rpath -Xlinker "$TORCH_LIB" \
  -o "$OUT/libref_kernels.so" "$HERE/ref_kernels_wrapper.cu"
echo "built $OUT/libref_kernels.so"
