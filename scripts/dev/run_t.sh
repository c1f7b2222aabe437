python -m pytest tests -x -q -m gpu > gpurun_out/pytest_t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_t.log
tail -3 gpurun_out/pytest_t.log
python scripts/small_grids.py --cases 2880x1440,3600x1800,4000x2000,4600x2300,5000x2500 --variants march,resident --steps 400 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d['case'],d['variant'],d.get('resident'),round(d.get('us_per_step',0),2),round(d.get('gupdates_per_s',0),1),d.get('bit_identical_to_first'),d.get('skipped','')[:60])
"
