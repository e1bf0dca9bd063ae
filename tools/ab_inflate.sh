# usage: bash tools/ab_inflate.sh [variants...] — device ingest time of builds of the C ABI with other look-ahead
# table widths (make -C sgcount_b200/csrc variant NAME=fast75 DEFS="-DSGC_INFLATE_FAST_BITS=7 -DSGC_INFLATE_DIST_FAST_BITS=5")
for v in "" "$@"; do
  echo "== variant ${v:-default}"
  SGC_CUDA_LIB=$PWD/sgcount_b200/lib/libsgcount_cuda${v:+_$v}.so DINF_SKIP_CLI=1 python tools/device_inflate_time.py 16777216 2>&1 | grep "device ingest" | tail -2 | cut -c1-160
done
