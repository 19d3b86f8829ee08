mkdir -p gpurun_out /tmp/tr
python - <<'PY'
import sys
sys.path.insert(0, '/root/repo')
import tests.test_gpu_two_ranks as t
w = t.WORKER % '/root/repo'
w = w.replace("'ratio': b.acceptRatio}", "'ratio': b.acceptRatio, 'rec': [{k: (np.asarray(v).tolist() if k != 'energies' else {a: np.asarray(x).tolist() for a, x in v.items()}) for k, v in r.items()} for r in b.walker_records]}")
open('/tmp/tr/worker.py', 'w').write(w)
PY
for v in ${VARIANTS:-base base base}; do
  rm -f /tmp/tr/result_*.json
  ( [ "$v" != "base" ] && export $v; MASTER_ADDR=127.0.0.1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29741 /tmp/tr/worker.py > /tmp/tr/out.log 2>&1 )
  python -c "
import json,glob
for f in sorted(glob.glob('/tmp/tr/result_*.json')):
    d=json.load(open(f)); print('$v', 'rank', d['rank'], 'local', d['local'], [[round(x,1) for x in w] for w in d['work']]); print('    rec0', {k: v for k, v in d['rec'][0].items()})
"
done
