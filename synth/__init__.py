"""Synthetic workloads: `opt` dictionaries, seeded feature generators and seeded checkpoints in the
reference's state_dict layout.  Pure data builders shared by bench.py, the tests and the oracle tooling;
nothing here touches the oracle or the reference."""
