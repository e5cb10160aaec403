timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
export EXP4="X=0|
PSA_LIB_PATH=$PWD/build/exp/libpsa_expS.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_bloom6.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_bloom8.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_bloom16.so|
X=1|"
bash scripts/gpu_exp4.sh
