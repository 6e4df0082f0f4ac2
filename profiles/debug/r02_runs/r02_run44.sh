timeout 300 python profiles/conv64_probe.py > gpurun_out/r02ar_conv64.txt 2>&1
cat gpurun_out/r02ar_conv64.txt | tail -6
