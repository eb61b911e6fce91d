"""sam_textvqa_b200 -- B200-native (sm_100a) implementation of the SA-M4C hot path of
yashkant/sam-textvqa: drop-in `SAM4C` module, spatial-graph builder, fused loss; all compute in
hand-written CUDA behind the C ABI of include/samk.h (libsamk.so)."""
