"""openrave_b200 — B200-native (sm_100a) batched collision checking for OpenRAVE's
"FK + environment + self collision per configuration" hot path.

Layout
  csrc/            CUDA kernels + the C ABI of include/cudacollision.h  -> libcudacollision.so
  checker.py       ctypes binding + CollisionCheckerBase-shaped host class (the plugin's harness twin)
  kinbody.py, environment.py, ravemath.py, xmlscene.py, ivloader.py, flatten.py, serialize.py
                   host-side mirror of the OpenRAVE objects the path reads (load-time only)
"""
