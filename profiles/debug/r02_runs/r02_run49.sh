timeout 300 python profiles/conv64_probe.py > gpurun_out/r02av_probe.txt 2>&1
tail -3 gpurun_out/r02av_probe.txt
