"""Oracle = CPU checker for the CSM + DAS-map hot path (test infrastructure, see restatement.py)."""
