"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/fluid_ref.h)."""
