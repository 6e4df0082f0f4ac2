timeout 600 python profiles/lanes_probe.py 1 2 3 > gpurun_out/r02ad_lanes.txt 2>&1
cat gpurun_out/r02ad_lanes.txt | tail -8
