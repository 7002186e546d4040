"""Drop-in mirror of the reference's ``Stage3_source`` package for the ProteoScribe sampling path.

Same module and function names as /root/reference/Stage3_source, backed by the CUDA engine.
``biom3_b200.install_as_stage3_source()`` registers these modules under the name
``Stage3_source`` so the reference's run_ProteoScribe_sample.py imports them unmodified."""
