python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02b_tests.txt
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02b_tests_all.txt
for p in 0 1 2 3 4 8; do echo "POLY=$p" >> gpurun_out/r02b_poly.txt; CRAFT_PV_POLY=$p timeout 120 python profiles/kernel_only.py pv 20 >> gpurun_out/r02b_poly.txt 2>&1; done
CRAFT_PV_POLY=3 python -m pytest tests -m gpu -q -k "attn_lse_pv or flow_matches" 2>&1 | tail -5 > gpurun_out/r02b_poly3_tests.txt
cat gpurun_out/r02b_tests.txt gpurun_out/r02b_poly.txt gpurun_out/r02b_poly3_tests.txt
