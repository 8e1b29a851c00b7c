import sys, ctypes as C, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import lcd_testlib as T
g=T.load_golden('wfa_utest')
p,t=g['pairs'][100]
par=T.WfaParams(*g['params']['affine.wfapt0'])
P=np.frombuffer(p.encode(),np.uint8); Tt=np.frombuffer(t.encode(),np.uint8)
if sys.argv[1]=='emu':
    emu=C.CDLL('tests/emu/libwfa_emu.so')
    ops=C.create_string_buffer(1000); res=T.WfaResult()
    emu.emu_wfa_align(P.ctypes.data_as(C.c_void_p),len(P),Tt.ctypes.data_as(C.c_void_p),len(Tt),C.byref(par),ops,C.byref(res),64)
    print(res.score)
else:
    import longcalld_b200 as lcd
    lcd.init()
    print(lcd.wfa_batch([(P,Tt)], tuple(g['params']['affine.wfapt0']))[0][:2])
