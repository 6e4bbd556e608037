import ctypes as C, sys
sys.path.insert(0,'.')
import simplexmethod_b200 as sm
L=sm.lib()
d=(C.c_double*2)()
for i in range(3):
    t=L.enumgpu_fp64_peak_detail(3,d); print(t,d[0],d[1], "derived rate", t*1e12/(2*32*148*d[1]*1e6) if d[1] else None)
