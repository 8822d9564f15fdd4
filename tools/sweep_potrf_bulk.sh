run() { echo "== $1"; env $1 python bench.py --quick --steps 24 --concurrent 4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value'],2), 'evals/s  potrf_ms', round(d['stage_ms']['potrf'],3), 'one-alone', round(d['roofline']['detail']['one_evaluation_alone']['ms'],3))"; }
run "GPRAS_X=0"
run "GPRAS_B200_BULK_L_RANK=512"
run "GPRAS_B200_PANEL_GROUP=8"
run "GPRAS_B200_PANEL_GROUP=8 GPRAS_B200_BULK_L_RANK=512"
run "GPRAS_B200_PANEL_GROUP=6 GPRAS_B200_BULK_L_RANK=512"
run "GPRAS_B200_PANEL_GROUP=8 GPRAS_B200_BULK_L_RANK=512 GPRAS_B200_PAIR_MIN_REM=16"
GPRAS_B200_BULK_L_RANK=512 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "potrf or building or lapack or golden" 2>&1 | tail -3
