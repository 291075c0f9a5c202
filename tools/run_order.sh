mkdir -p gpurun_out; O=gpurun_out/r2_v43
B="python bench.py --no-cpu-baseline --no-e2e --workload cfg5"
for ord in row tile tile_theta; do for ar in "" "--as-rank 3/8"; do
$B --pose-order $ord $ar > ${O}_tmp.json 2>>${O}_order.err; python - "$ord" "$ar" <<'PY' | tee -a gpurun_out/r2_v43_order.txt
import json,sys
d=json.loads(open("gpurun_out/r2_v43_tmp.json").read().strip().splitlines()[-1])
print(sys.argv[1], sys.argv[2] or "N=1", "ms", round(d["ms_per_step"],4), [round(x,4) for x in d["reps_ms_per_step"]], "clean ms", round(d["clean_map"]["ms_per_step"],4), "list ms", round(d["list_order"]["ms_per_step"],4), "parity", d["parity"]["pass_rate_vs_shifted_reference"])
PY
done; done
