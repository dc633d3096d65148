"""Host emulation of the device-logic backend (test infrastructure; never linked into libcask_b200.so)."""
