# code-size variants of the guide-site kernel (TQ_SITE_NOINLINE bits, TQ_PHILOX_UNROLL) on one 8-GPU rank's C3 shard:
# per-kernel times at the initial point and after 1000 / 3000 SVI iterations
for v in "$@"; do
  LIBV=""; [ "$v" != "default" ] && LIBV="$PWD/tapqir_b200/lib/libtapqir_b200.$v.so"
  echo "== $v"; TQ_LIB=$LIBV timeout 300 python profiles/kernel_times_trained.py c3s8 3000 2>&1 | grep -E "^---" | sed -E 's/.*(initial point|after [0-9]+ iterations).*(site_fast_kernel<[0-9]+> [0-9]+).*(sum [0-9]+ us).*/\1: \2; \3/'
done
