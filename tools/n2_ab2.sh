# N=2 variants of the LM step with the per-category breakdown (run under `gpurun --gpus 2`)
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=2 $*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['roofline']['breakdown_ms'].items()}, d['clocks']['sm_mhz'], d['config'].get('dp_comm'))"
}
python bench.py --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['roofline']['breakdown_ms'].items()}, d['clocks']['sm_mhz'])"
run A=1
run SK_BENCH_NO_HOSTSUM=1
run SK_P2P_CTAS=16
run SK_DP_COMM=nccl SK_BENCH_NO_HOSTSUM=1
