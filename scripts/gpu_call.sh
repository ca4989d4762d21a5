timeout 300 python -m pytest tests/test_gpu_attention.py -x -q --timeout 60 --timeout-method=thread 2>&1 | grep -v "^  warn\|Warning" | tail -15
timeout 300 python -m pytest tests/test_gpu_model.py -x -q --timeout 120 --timeout-method=thread 2>&1 | tail -4
