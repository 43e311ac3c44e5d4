cd /root/repo
for cfg in "--contigs-per-gpu 8 --sync spin" "--contigs-per-gpu 8 --sync yield" "--contigs-per-gpu 16 --sync yield" "--contigs-per-gpu 12 --sync yield"; do
  python bench.py --no-cpu-baseline --no-other-paths $cfg > /tmp/o.json 2>/tmp/o.err || { tail -3 /tmp/o.err; continue; }
  python -c "
import json; j=json.load(open('/tmp/o.json')); print('$cfg', 'value %.1fM step %.2f ms e2e %.2fM  %s' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, {k:round(v,2) for k,v in j['stage_ms'].items() if k.startswith('wall') or k.startswith('host')}))"
done
