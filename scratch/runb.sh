python bench.py --steps 400 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/b1.json')); r=d['roofline']; print('step us', d['ms_per_step']*1e3, 'train us', r['us_per_launch'], 'frac', r['frac'], 'post us', r['postprocess_kernel']['us_per_launch'], 'step frac', r['step']['frac'])"; tail -3 gpurun_out/b1.err
