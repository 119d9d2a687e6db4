# The reference's backend switch (utils/interpol/backend.py:1).  Here the CUDA library IS the backend.
jitfields = False
cuda_native = True
