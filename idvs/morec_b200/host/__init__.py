"""Host side of the drop-in packages (`inbatch_sasrec_e2e_text/`, `inbatch_sasrec_e2e_vision/` at the repo root):
command line, data preparation, training loop, evaluation and checkpoints with the reference's file names, function
names, arguments and observable behaviour (log lines, `epoch-N.pt` layout), re-implemented around the morec_b200
kernels -- device-resident item content and batch assembly, fused optimizer + GradScaler protocol
(N3), full-catalogue rank kernel (N1), reference-compatible checkpoints (N4)."""
