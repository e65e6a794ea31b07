"""BoardGameEnv façades (reference: alpha_zero/envs/) whose state lives in engine slots on the GPU."""
