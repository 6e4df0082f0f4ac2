t=r02as
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c64 -c 1 -f -o gpurun_out/${t}_conv python profiles/conv64_probe.py > gpurun_out/${t}_ncu.log 2>&1
python profiles/ncu_key_metrics.py gpurun_out/${t}_conv.ncu-rep gpurun_out/${t}_conv_traffic.json > gpurun_out/${t}_conv_ncu_full.txt 2>&1
ncu -i gpurun_out/${t}_conv.ncu-rep --page source --csv --print-source sass > gpurun_out/_sass.csv 2>/dev/null
python profiles/ncu_sass_stalls.py gpurun_out/_sass.csv 40 > gpurun_out/${t}_conv_sass_stalls.txt 2>&1
rm -f gpurun_out/${t}_conv.ncu-rep gpurun_out/_sass.csv
cat gpurun_out/${t}_conv_ncu_full.txt gpurun_out/${t}_conv_sass_stalls.txt
