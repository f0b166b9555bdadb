# split guide-site passes as the default: exact parameter checksum (must equal the one-launch form's, r2s2_site_split_ab.sh),
# per-kernel times, the default bench line, the GPU suite
mkdir -p gpurun_out
timeout 120 python profiles/r2s2_determinism.py c3s8 1000 500 1 2>&1 | grep "^iter"
timeout 200 python profiles/kernel_times_trained.py c3s8 3000 2>&1 | grep -E "^---" | sed -E 's/ksmogn_stream_kernel<unsigne [0-9]+; //; s/adam_kernel[^;]*; //; s/globals_[a-z_<>]* [0-9]+; //g'
timeout 400 python bench.py > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/final3_bench.json").read().strip().splitlines()[-1])
print("C3", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,1), "M | trained", d["trained_state"], "| launches", d["gpu_launches"], "| loss", d["final_loss"])
for k,v in d["sub_results"].items(): print(k, round(v["ms_per_step"],4), round(v["value"]/1e6,1), v.get("trained_state"))
P
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/final3_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final3_tests.log
