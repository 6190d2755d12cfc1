mkdir -p gpurun_out
export IIC_B200_XCHG_TIMEOUT_MS=30000
run() { # n cfg tag steps
  n=$1; c=$2; tag=$3; st=$4
  if [ "$n" = "1" ]; then timeout 300 python bench.py --gpus 1 --config $c --steps $st --warmup 5 --no-extra > gpurun_out/r2j_${tag}.json 2> gpurun_out/r2j_${tag}.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n+c*10)) bench.py --gpus $n --config $c --steps $st --warmup 5 --no-extra > gpurun_out/r2j_${tag}.json 2> gpurun_out/r2j_${tag}.err; fi
  python - <<PY
import json
raw=open("gpurun_out/r2j_${tag}.json").read()
try:
    d=json.loads(raw[raw.index('{"metric"'):].splitlines()[0])
    print("${tag}", d["n_gpus"], d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["multi_gpu_check"], d["gpu_launches_per_step"])
except Exception as e:
    print("${tag} FAILED", e)
PY
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dp_parity_check.py > gpurun_out/r2j_dp8.txt 2>&1; grep "DP PARITY\|FAIL" gpurun_out/r2j_dp8.txt | head
run 1 2 c2n1 500; run 2 2 c2n2 500; run 4 2 c2n4 500; run 8 2 c2n8 500
run 8 3 c3n8 50; run 1 3 c3n1 50
run 2 4 c4n2 20; run 8 4 c4n8 20
run 8 5 c5n8 20
