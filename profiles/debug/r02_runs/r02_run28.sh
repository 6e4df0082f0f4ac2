t=r02ab
KO_PROFILE=1 CRAFT_B200_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:scores_kernel -c 1 -f -o gpurun_out/${t}_corr python profiles/kernel_only.py corr 1 > gpurun_out/${t}_ncu_corr.log 2>&1
python profiles/ncu_key_metrics.py gpurun_out/${t}_corr.ncu-rep gpurun_out/${t}_corr_traffic.json > gpurun_out/${t}_corr_ncu_full.txt 2>&1
cat gpurun_out/${t}_corr_ncu_full.txt | head -40
