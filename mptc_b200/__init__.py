"""mptc_b200 -- B200-native MPTC encoder hot path (DXT1 fit + index-reuse search +
endpoint wavelet) behind a C-ABI.  See DESIGN.md / INTEGRATION.md."""
