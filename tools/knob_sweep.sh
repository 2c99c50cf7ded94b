run() { echo "== $*"; env "$@" timeout 120 python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['value'])"; }
run A=1
run GSN_STREAM_TARGET_US=0.9
run GSN_STREAM_TARGET_US=1.1
run GSN_STREAM_TARGET_US=1.2
run GSN_POLL_NS=50
run GSN_POLL_NS=200
run GSN_STREAM_SMS=144
run GSN_STREAM_SMS=136
run GSN_XOP_RING=32
run GSN_XOP_RING=128
run A=2
