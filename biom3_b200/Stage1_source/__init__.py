"""Mirror of the one Stage-2 piece the sampling path consumes: the Facilitator (z_t -> z_c),
/root/reference/Stage1_source/model.py:473-493 (BASELINE.json configs[3])."""
