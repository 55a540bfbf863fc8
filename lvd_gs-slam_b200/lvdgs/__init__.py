"""Host-side plumbing of the B200 rasterizer: native library loader, synthetic scenes, keyframe-sharded mapping."""
