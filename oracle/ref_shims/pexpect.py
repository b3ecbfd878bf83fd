"""Stand-in for pexpect: the particle simulator binary cannot run here; importing is enough."""
