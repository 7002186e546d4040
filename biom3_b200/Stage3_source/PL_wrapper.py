"""Import shim for /root/reference/Stage3_source/PL_wrapper.py.

The reference CLI imports this module (run_ProteoScribe_sample.py:10) but never uses it on the
sampling path; the Lightning/DeepSpeed training wrapper it defines is out of scope here."""


class PL_ProtARDM:  # noqa: N801 - reference class name
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('training (PL_ProtARDM) is outside the biom3_b200 sampling path')
