mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r4e_pytest.txt
bash tools/s8_ab.sh r4e_ab.txt k4b128 - > /dev/null
cat gpurun_out/r4e_pytest.txt gpurun_out/r4e_ab.txt
