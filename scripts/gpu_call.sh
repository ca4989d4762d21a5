timeout 300 python -m pytest tests/test_gpu_model.py -x -q --timeout 120 --timeout-method=thread 2>&1 | grep -E "^E |assert|Error" | head -12
