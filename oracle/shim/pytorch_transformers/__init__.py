"""TEST INFRASTRUCTURE ONLY (oracle shim) -- see modeling_bert.py."""
