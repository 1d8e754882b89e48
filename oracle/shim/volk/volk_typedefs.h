/* oracle shim: intentionally empty (reference includes it from cc_encoder.cpp) */
