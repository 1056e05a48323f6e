"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch fp32 on the host) of the reference's denoising hot path
(rafaelsoStanford/State_Policy_DiffusionModel).  Nothing in the product package
(`state_policy_diffusionmodel_b200/`) imports this; only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may.

Parity status
  * U-Net / encoder / wrapper restatement (`unet_ref.py`, `sampler_ref.py`): PINNED against the
    reference's own modules imported from /root/reference in the build container
    (`oracle/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).
  * Scheduler arithmetic (`schedulers.py`): the source (diffusers==0.17.1, requirements.txt:30) is
    NOT vendored in the reference and not installable here => "parity unpinned" against the real
    package; pinned only against SURVEY.md A.4 known-answer values and self-consistency identities.
"""
