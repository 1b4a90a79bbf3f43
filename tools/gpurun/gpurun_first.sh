set -x
nvidia-smi --query-gpu=name,driver_version --format=csv
python -c "import similari_b200.engine as e; from similari_b200._lib import lib; print('devices', lib().sb200_device_count())"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40
