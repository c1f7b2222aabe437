python -m pytest tests -x -q -m gpu > gpurun_out/pytest_t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_t.log
tail -3 gpurun_out/pytest_t.log
python scripts/small_grids.py --cases 360x180,720x360d,1440x720,1440x720d,2000x1000d --variants march,resident --steps 1000 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d['case'],d['variant'],d.get('resident'),round(d.get('us_per_step',0),2),d.get('bit_identical_to_first'),d.get('skipped','')[:60])
"
