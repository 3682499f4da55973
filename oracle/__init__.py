"""Parity oracle (test infrastructure).  See ps_vae_oracle.py / philox_ref.py headers.  Never imported by the product package."""
