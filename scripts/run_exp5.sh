timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
export EXP4="X=0|
X=1|"
bash scripts/gpu_exp4.sh
