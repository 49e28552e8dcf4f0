"""ORACLE / TEST INFRASTRUCTURE ONLY (spl stand-in)."""
