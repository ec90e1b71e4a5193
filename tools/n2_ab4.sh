one() { CUDA_VISIBLE_DEVICES=$1 python bench.py --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 gpu$1', round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['roofline']['breakdown_ms'].items()}, d['clocks']['sm_mhz'])"; }
one 0
one 1
(one 0 &) ; one 1; sleep 5
env python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,2) for k,v in d['roofline']['breakdown_ms'].items()}, d['clocks']['sm_mhz'], d['config'].get('dp_comm'))"
