"""Drop-in twin of the prover side of verifiable_mpc.trinocchio.pynocchio with the BN256 MSMs on the B200."""
