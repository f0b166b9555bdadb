# site kernel variants on a TRAINED model: per-kernel times at the initial point and after 1000 / 3000 iterations
for v in "$@"; do
  LIBV=""; [ "$v" != "default" ] && LIBV="$PWD/tapqir_b200/lib/libtapqir_b200.$v.so"
  echo "== $v"; TQ_LIB=$LIBV timeout 600 python profiles/kernel_times_trained.py c3s8 3000 2>&1 | grep -E "^---" | sed 's/ksmogn_stream_kernel<unsigne [0-9]*; //'
  TQ_LIB=$LIBV timeout 600 python profiles/kernel_times_trained.py c2 3000 2>&1 | grep -E "^---" | tail -1
done
