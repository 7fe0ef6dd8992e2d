"""Drop-in twins of verifiable_mpc.ac20.{pivot, compressed_pivot} with the group arithmetic on the B200."""
