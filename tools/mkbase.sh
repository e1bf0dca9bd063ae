# usage: bash tools/mkbase.sh [rev] [out.so] — builds the C ABI of a committed revision (default HEAD)
# into sgcount_b200/lib_base.so, the baseline of tools/ab.sh
rev=${1:-HEAD}; out=${2:-sgcount_b200/lib_base.so}
tmp=$(mktemp -d) && git archive "$rev" sgcount_b200/csrc include | tar -x -C "$tmp" &&
  make -C "$tmp/sgcount_b200/csrc" >/dev/null 2>&1 && cp "$tmp/sgcount_b200/lib/libsgcount_cuda.so" "$out" && echo "built $rev -> $out"
rm -rf "$tmp"
