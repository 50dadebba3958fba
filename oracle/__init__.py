"""CPU oracle for the stereo hot path -- TEST INFRASTRUCTURE ONLY (see stereo_oracle.c header)."""
