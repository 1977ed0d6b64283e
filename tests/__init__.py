"""Parity and host-logic tests of align3r_b200 (GPU tests carry the `gpu` marker)."""
