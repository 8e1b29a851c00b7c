import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import longcalld_b200 as lcd
from bench import Workload, _vp
from longcalld_b200.capi import _phase_structs, EDLIB_RESULT_DTYPE, POA_PARAMS_DTYPE, POA_RESULT_DTYPE, WFA_PARAMS_DTYPE, WFA_RESULT_DTYPE
lcd.init(0, 0)
wl = Workload(50, "hifi", 11)
L = lcd.lib()
ins, outs, keep, res = _phase_structs(wl.phase, -9)
for _ in range(3):
    t0 = time.perf_counter(); rc = L.lcd_phase_batch(C.c_int(len(wl.phase)), ins, outs); t1 = time.perf_counter()
    print("phase_batch %.1f ms rc=%d" % ((t1 - t0) * 1e3, rc))
eseqs, eqo, eql, eto, etl = wl.edlib
ne = len(eql)
emode = np.zeros(ne, np.int32); ewant = np.ones(ne, np.int32); eres = np.zeros(ne, dtype=EDLIB_RESULT_DTYPE)
eoff = np.zeros(ne + 1, dtype=np.int64); np.cumsum(eql.astype(np.int64) + etl + 2, out=eoff[1:])
ealn = np.zeros(int(eoff[-1]) + 1, dtype=np.uint8)
for _ in range(3):
    t0 = time.perf_counter(); rc = L.lcd_edlib_batch(C.c_int(ne), _vp(eseqs), C.c_size_t(eseqs.size), _vp(eqo), _vp(eql), _vp(eto), _vp(etl), _vp(emode), _vp(ewant), _vp(ealn), _vp(eoff), _vp(eres)); t1 = time.perf_counter()
    print("edlib_batch %.1f ms rc=%d" % ((t1 - t0) * 1e3, rc))
n = wl.n_poa
ppar = np.zeros(n, dtype=POA_PARAMS_DTYPE); ppar[:] = lcd.poa_params()
buf, R, ref_off, ref_len, txt_off = wl.wfa_layout(); cons = buf[R:]
pres = np.zeros(n, dtype=POA_RESULT_DTYPE)
for _ in range(3):
    t0 = time.perf_counter()
    rc = L.lcd_poa_batch(C.c_int(n), _vp(wl.seqs), C.c_size_t(wl.seqs.size), _vp(wl.first), _vp(wl.n_reads), _vp(wl.read_off), _vp(wl.read_len), C.c_int(len(wl.read_len)), _vp(ppar), _vp(cons), _vp(wl.cons_off), None, None, None, _vp(pres))
    t1 = time.perf_counter(); print("poa_batch %.1f ms rc=%d" % ((t1 - t0) * 1e3, rc))
wpar = np.zeros(n, dtype=WFA_PARAMS_DTYPE); wpar[:] = lcd.wfa_params(); wres = np.zeros(n, dtype=WFA_RESULT_DTYPE)
for _ in range(3):
    t0 = time.perf_counter()
    tl = np.ascontiguousarray(pres["cons_len"]); cap = 2 * (ref_len.astype(np.int64) + tl) + 8
    off = np.zeros(n + 1, dtype=np.int64); np.cumsum(cap, out=off[1:]); ops = np.empty(int(off[-1]) + 1, dtype=np.uint8)
    rc = L.lcd_wfa_batch(C.c_int(n), _vp(buf), C.c_size_t(buf.size), _vp(ref_off), _vp(ref_len), _vp(txt_off), _vp(tl), _vp(wpar), ops.ctypes.data_as(C.c_char_p), _vp(off), _vp(wres))
    t1 = time.perf_counter(); print("wfa_batch %.1f ms rc=%d" % ((t1 - t0) * 1e3, rc))
