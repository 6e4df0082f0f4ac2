t=r02ai
timeout 1400 python -m pytest tests -m gpu -x -q --timeout=600 2>&1 | tail -8 > gpurun_out/${t}_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/${t}_tests.txt 2>&1
cat gpurun_out/${t}_tests.txt | tail -12
