mkdir -p gpurun_out
python - <<'PY' | tee gpurun_out/gather.txt
import importlib
psa = importlib.import_module("rust-pseudoaligner_b200.pseudoaligner")
for tb in (1 << 26, 1 << 29, 600 << 20, 1 << 31, 1 << 33):
    for cb in (0, 32, 64, 128):
        print("table %5d MB chunk %3d B: %.0f GB/s" % (tb >> 20, cb, psa.gather_probe(0, tb, cb, 64)))
PY
