"""Shim for `vtp.models` (reference: vtp/models/__init__.py): only `vtp_hf` is provided; see compat/vtp/__init__.py."""
