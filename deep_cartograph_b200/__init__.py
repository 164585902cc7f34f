"""deep_cartograph_b200 -- B200-native drop-in for deep_cartograph's data-parallel CV hot path.

(frames x features) matrix -> per-feature standardisation -> C0 / C_tau accumulation
(PCA / TICA / hTICA / DeepTICA loss) -> F x F generalised eigenproblem -> projection of every
frame -> KMeans assign / update in CV space, behind the reference's ``train_colvars`` /
``traj_cluster`` step interfaces, CV-calculator classes and YAML configuration.

Device code: hand-written sm_100a CUDA in ``csrc/`` behind the C-ABI of ``include/dcg.h``.
There is no CPU fallback; the CPU restatement in ``oracle/`` is test infrastructure only.
"""
__version__ = "0.1.0"
