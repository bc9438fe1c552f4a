"""CPU oracle for the simulation hot path -- test infrastructure, never imported by unitaria_b200."""
