"""Mirror of /root/reference/Stage3_source/transformer_sampling_helper.py, which holds imports only
(:1-9, no functions).  The sampling helpers live in sampling_analysis / transformer_training_helper."""
