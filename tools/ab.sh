# usage: bash tools/ab.sh <baseline.so> [tune configs...]  — same-box A/B of two builds of the C ABI
base=$1; shift
for rep in 1 2; do
  echo "== baseline $base"; SGC_CUDA_LIB=$PWD/$base python tools/tune.py "$@" 2>&1 | grep warps
  echo "== current"; python tools/tune.py "$@" 2>&1 | grep warps
done
