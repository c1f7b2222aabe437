# dev probe: what the CFL estimate (4 FP64 instructions per cell) costs the step kernel -- march_kernel's own duration (ncu launch list)
# with the default library and with one built with -DSWB_NO_CFL_EST (no estimate: the fix-up pass then runs every step, timed separately)
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
for lib in libswb200 libswb200_noest; do
SWB_LIB=sw_solver_b200/$lib.so ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3_cflest_$lib.csv $CMD > gpurun_out/r3_cflest_$lib.log 2>&1; echo "$lib rc=$?"
python - "$lib" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/r3_cflest_{sys.argv[1]}.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
t = [float(r[v].replace(",", "")) for r in rows[1:] if "march_kernel" in r[k]]
print(sys.argv[1], "march_kernel launches", len(t), "mean of the last 5: %.1f us" % (sum(t[-5:]) / 5 / 1e3 if t and t[0] > 1e5 else sum(t[-5:]) / 5))
PY
done
