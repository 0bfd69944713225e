"""scd_b200 - B200-native (sm_100a) clustering-and-naming inner loop of Visual-AI/SCD.

Drop-in surface (reference names kept):
  * ``scd_b200.kmeans.K_Means``            <- local_utils/faster_mix_k_means_pytorch.py:8 and
                                              gcd/methods/clustering/faster_mix_k_means_pytorch.py:47 (``mode=``)
  * ``scd_b200.kmeans.pairwise_distance``  <- local_utils/faster_mix_k_means_pytorch.py:177
  * ``scd_b200.naming.score_topk`` / ``clip_preds`` / ``vote`` / ``assign_name`` / ``reassign`` and the
    two voting loops                       <- main_unsup.py:504-614, main_ptsup.py:526-676,
                                              local_utils/clip_lang_util.py:151-180
All device work goes through ``libscd_b200.so`` (``include/scd_b200.h``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['kmeans', 'naming', 'dist', 'synth']
