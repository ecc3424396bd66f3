"""CPU oracle for the MIPHEI-ViT hot path — TEST INFRASTRUCTURE ONLY (see oracle/model.py header)."""
