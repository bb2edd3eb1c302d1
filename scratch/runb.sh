for mode in "" "--two-streams"; do
python bench.py --steps 800 --warmup 20 --no-e2e --no-cpu-baseline $mode > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/b1.json')); r=d['roofline']; print('$mode step us %.2f value %.3fM train us %.2f frac %.3f post us %.2f step frac %.3f' % (d['ms_per_step']*1e3, d['value']/1e6, r['us_per_launch'], r['frac'], r['postprocess_kernel']['us_per_launch'], r['step']['frac']))"; tail -3 gpurun_out/b1.err
done
