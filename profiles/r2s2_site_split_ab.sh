# the guide-site kernel's two passes as two launches (TQ_SITE_SPLIT=1) against one launch: exact parameter checksum after
# 1000 iterations of one 8-GPU rank's C3 shard (must be identical), then per-kernel times at the initial point / 1000 / 3000
for sp in 0 1; do
  echo "== TQ_SITE_SPLIT=$sp"
  TQ_SITE_SPLIT=$sp timeout 120 python profiles/r2s2_determinism.py c3s8 1000 500 1 2>&1 | grep "^iter"
  TQ_SITE_SPLIT=$sp timeout 200 python profiles/kernel_times_trained.py c3s8 3000 2>&1 | grep -E "^---" | sed -E 's/ksmogn_stream_kernel<unsigne [0-9]+; //; s/adam_kernel[^;]*; //; s/globals_[a-z_<>]* [0-9]+; //g'
done
