python tools/prof_labels.py > gpurun_out/plain_labels.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_labels.csv python tools/prof_labels.py > gpurun_out/ncu_l_labels.log 2>&1
grep -v "^==" gpurun_out/launches_labels.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size')
for r in rows[-16:]:
    if len(r)>iv: print(r[0], r[ik][:80], r[ig], r[iv])
"
