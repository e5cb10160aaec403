# A/B harness of one experiment batch (results: profiles/r1_exp_overlap_knobs.txt).  The libpsa_*.so variants under build/exp/
# are built first with: make -C rust-pseudoaligner_b200/csrc OUT=../../build/exp/libpsa_<name>.so NVFLAGS="<flags> -D<switch>=<value>" ../../build/exp/libpsa_<name>.so
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
export EXP4="X=0|
X=1|"
bash scripts/gpu_exp4.sh
