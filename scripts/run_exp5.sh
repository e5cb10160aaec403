export EXP4="X=0|
PSA_LIB_PATH=$PWD/build/exp/libpsa_pfsucc.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_pfspan.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_pfboth.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_pfboth32.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_wi0.so|
PSA_LIB_PATH=$PWD/build/exp/libpsa_wi6.so|"
bash scripts/gpu_exp4.sh
