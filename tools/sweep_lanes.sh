for L in 1 2 4 8 32; do echo "LANES=$L"; GBENV_LANES=$L timeout 300 python bench.py --steps 20 --warmup 3 --cpu-baseline-seconds 1 --e2e-steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['emulated_instr_per_s']/1e9, d['e2e']['value'])
"; done
echo "N=32768"; for L in 8 16 32; do GBENV_LANES=$L timeout 300 python bench.py --envs-per-gpu 32768 --steps 10 --warmup 3 --cpu-baseline-seconds 1 --e2e-steps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['emulated_instr_per_s']/1e9, d['e2e']['value'])
"; done
