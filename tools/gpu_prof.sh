#!/bin/bash
# profiles of the final build: ncu launch list of the full step, full captures of the row-team decoder (fwd + bwd) and
# of the stem convolution.  Every command first runs WITHOUT ncu and must exit 0.
O=gpurun_out/${1:-prof}; mkdir -p $O
timeout 300 python tools/full_step.py --steps 2 > $O/full_step.log 2>&1 || { echo "full_step failed"; tail -5 $O/full_step.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches.csv python tools/full_step.py --steps 2 > $O/ncu_launches.log 2>&1
echo "launch list rc=$?" > $O/rc.txt
python tools/ncu_extract.py launches $O/launches.csv 2 > $O/launches_full_step.md; head -30 $O/launches_full_step.md
gzip -f $O/launches.csv
timeout 300 python tools/head_step.py --steps 2 > $O/head_step.log 2>&1 || { echo "head_step failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_team -s 2 -c 2 -o $O/team_full -f python tools/head_step.py --steps 2 > $O/ncu_team.log 2>&1
echo "ncu team rc=$?" >> $O/rc.txt
python tools/ncu_extract.py full $O/team_full.ncu-rep > $O/ncu_full_decode_team.csv; wc -l $O/ncu_full_decode_team.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_conv -s 1 -c 1 -o $O/stem_full -f python tools/stem_probe.py > $O/ncu_stem.log 2>&1
echo "ncu stem rc=$?" >> $O/rc.txt
python tools/ncu_extract.py full $O/stem_full.ncu-rep > $O/ncu_full_stem_conv.csv; wc -l $O/ncu_full_stem_conv.csv
ls -la $O; cat $O/rc.txt
