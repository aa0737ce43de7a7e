"""Import shim (test infrastructure only): lets /root/reference import on a box without apex."""
