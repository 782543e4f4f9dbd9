import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); e = d.get("e2e") or {}; sw = d.get("stats_window") or {}
            print(f"{f.split('/')[-1][:-5]:32s} n={d['n_gpus']} value {d['value']/1e9:7.2f} G  {d['ms_per_step']*1e3:6.2f} us  e2e {e.get('value', 0)/1e9:6.2f} G ref-wire {e.get('value_reference_wire', 0)/1e9:6.2f} G  allreduce_ok={sw.get('allreduce_equals_sum_of_shards')} {sw.get('allreduce_us_avg50')} stats={d.get('stats')}")
