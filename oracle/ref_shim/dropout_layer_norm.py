"""Import shim (test infrastructure only): flash_attn.ops.rms_norm imports this extension."""
