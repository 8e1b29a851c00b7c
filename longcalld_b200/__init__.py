"""longcalld_b200 -- B200 (sm_100a) implementation of the re-alignment hot path of `longcallD call`.

The product is the C-ABI shared library ``longcalld_b200/csrc/liblcd_gpu.so`` (declared in
``include/lcd_gpu.h``); this package is the thin Python host binding used by the tests, bench.py
and the multi-GPU driver.  There is no CPU fallback: importing works anywhere, but every compute
call raises ``LcdGpuError`` when the library or a B200 is missing.
"""
from .capi import (LcdGpuError, lib, lib_path, build_library, init, shutdown, launch_count, reserve_sms, reserve_plan_memory, split_pool, stream, aux_stream, set_thread_stream,  # noqa: F401
                   WfaParams, WfaResult, wfa_params, WfaPlan, wfa_batch,
                   HEUR_NONE, HEUR_ADAPTIVE, HEUR_ZDROP,
                   PoaParams, poa_params, PoaPlan, poa_batch, poa_ncons_batch, pack_poa,
                   MODE_NW, MODE_SHW, MODE_HW, EdlibPlan, edlib_batch, xgaps,
                   PhasePlan, phase_batch, PileupPlan, pileup_batch, profile_batch, DigarPlan, digar_batch, digar_md_batch, digar_tags_batch, PileupOnDigarPlan, ProfileOnDigarPlan,
                   SitesPlan, sites_batch, PileupOnSitesPlan, ClassifyPlan, classify_batch, ClassifyOnPileupPlan, NoisyRegPlan, NoisyRegOnClassifyPlan, noisyreg_batch, SdustPlan, sdust_batch)

__all__ = ["LcdGpuError", "lib", "lib_path", "build_library", "init", "shutdown", "launch_count", "reserve_sms", "reserve_plan_memory", "split_pool", "stream", "aux_stream", "set_thread_stream",
           "WfaParams", "WfaResult", "wfa_params", "WfaPlan", "wfa_batch",
           "HEUR_NONE", "HEUR_ADAPTIVE", "HEUR_ZDROP",
           "PoaParams", "poa_params", "PoaPlan", "poa_batch", "poa_ncons_batch", "pack_poa",
           "MODE_NW", "MODE_SHW", "MODE_HW", "EdlibPlan", "edlib_batch", "xgaps",
           "PhasePlan", "phase_batch", "PileupPlan", "pileup_batch", "profile_batch", "DigarPlan", "digar_batch", "digar_md_batch", "digar_tags_batch", "PileupOnDigarPlan", "ProfileOnDigarPlan",
           "SitesPlan", "sites_batch", "PileupOnSitesPlan", "ClassifyPlan", "classify_batch", "ClassifyOnPileupPlan", "NoisyRegPlan", "NoisyRegOnClassifyPlan", "noisyreg_batch", "SdustPlan", "sdust_batch"]
