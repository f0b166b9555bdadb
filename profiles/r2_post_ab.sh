# A/B of local_post occupancy variants (csrc/build.py --variant=pbN -DTQ_POST_MINB=N): step time and per-kernel times
mkdir -p gpurun_out
for v in "$@"; do
  LIBV=""; [ "$v" != "default" ] && LIBV="$PWD/tapqir_b200/lib/libtapqir_b200.$v.so"
  for w in c3 c2; do
    echo "== $v $w"; TQ_LIB=$LIBV timeout 600 python profiles/kernel_times.py $w 10 2>&1 | grep -E "local_post|sum of kernels|adam_kernel<float"
  done
done
