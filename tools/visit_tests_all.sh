cd /root/repo; O=gpurun_out/r2ab; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=line > $O/pytest_all.log 2>&1; echo "pytest exit $?"
tail -n 60 $O/pytest_all.log
