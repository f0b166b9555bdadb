# A/B of build variants by per-kernel times (torch profiler): bash profiles/r2_post_ab.sh "<workloads>" variant...
mkdir -p gpurun_out
WL=$1; shift
for v in "$@"; do
  LIBV=""; [ "$v" != "default" ] && LIBV="$PWD/tapqir_b200/lib/libtapqir_b200.$v.so"
  for w in $WL; do
    echo "== $v $w"; TQ_LIB=$LIBV timeout 600 python profiles/kernel_times.py $w 10 2>&1 | grep -E "local_post|site_fast|sum of kernels"
  done
done
