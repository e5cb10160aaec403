# A/B harness of one experiment batch (results: profiles/r1_exp_overlap_knobs.txt).  The libpsa_*.so variants under build/exp/
# are built first with: make -C rust-pseudoaligner_b200/csrc OUT=../../build/exp/libpsa_<name>.so NVFLAGS="<flags> -D<switch>=<value>" ../../build/exp/libpsa_<name>.so
export EXP4="X=0|--value-mappers 1
X=0|--value-mappers 2
X=0|--value-mappers 3
PSA_L2_FETCH=32|--value-mappers 1
PSA_L2_FETCH=128|--value-mappers 1
PSA_LIB_PATH=$PWD/build/exp/libpsa_cs.so|--value-mappers 1
PSA_LIB_PATH=$PWD/build/exp/libpsa_mb9.so|--value-mappers 1
PSA_LIB_PATH=$PWD/build/exp/libpsa_mb9.so|--value-mappers 2"
bash scripts/gpu_exp4.sh
