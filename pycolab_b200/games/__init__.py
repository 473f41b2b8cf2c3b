"""Game definitions for the lowered example programs (set-up code only)."""
