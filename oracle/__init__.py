"""CPU oracle of the attention hot path — TEST INFRASTRUCTURE ONLY (see attention_oracle.py)."""
