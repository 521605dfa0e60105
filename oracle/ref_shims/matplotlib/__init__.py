"""Stub matplotlib: just enough names for the reference's unconditional GUI
imports (cube_env.py:4-5,7; cube_interactive.py:12-13,224; utils.py:9;
cube.py:2-4) to succeed.  Test-only; nothing is ever drawn."""
