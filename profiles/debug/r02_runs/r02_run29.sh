t=r02ac
for k in pv gru_zr lse; do
  case $k in pv) rx=attn_pv_kernel; n=1;; gru_zr) rx=shift_gemm_kernel; n=2;; lse) rx=scores_kernel; n=1;; esac
  KO_PROFILE=1 CRAFT_B200_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:$rx -c $n -f -o gpurun_out/${t}_$k python profiles/kernel_only.py $k 1 > gpurun_out/${t}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${t}_*
