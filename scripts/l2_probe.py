"""Development tool: L2 persistence limits of the device (cudaDevAttrMaxPersistingL2CacheSize = 108, cudaDevAttrMaxAccessPolicyWindowSize = 109, cudaDevAttrL2CacheSize = 38)."""
import ctypes as C, json
import torch
torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
out = {}
for name, a in (("l2_bytes", 38), ("max_persisting_l2_bytes", 108), ("max_access_policy_window_bytes", 109)):
    v = C.c_int(0); rt.cudaDeviceGetAttribute(C.byref(v), a, 0); out[name] = v.value
print(json.dumps(out))
