"""B200-native PhysicsNet training step (drop-in for Luka140/paig_reproduction's PhysicsNet hot path)."""
