export LD_LIBRARY_PATH=$(python - <<'PY'
import os
try:
    import nvidia.cuda_runtime as m
    print(os.path.join(list(m.__path__)[0], "lib"))
except Exception:
    print("/usr/local/cuda/lib64")
PY
):/usr/local/cuda/lib64:${LD_LIBRARY_PATH:-}
