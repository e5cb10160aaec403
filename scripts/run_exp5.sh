export EXP4="X=0|
PSA_TILE=1|
PSA_LIB_PATH=$PWD/build/exp/libpsa_rlast.so|"
bash scripts/gpu_exp4.sh
