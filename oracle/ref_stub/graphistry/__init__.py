"""Empty stand-in: the reference imports graphistry for (commented-out) debugging only."""
