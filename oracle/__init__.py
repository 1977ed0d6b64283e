"""CPU oracle of the global-alignment path: TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and the cpu_baseline / reference legs of bench.py; never by align3r_b200)."""
