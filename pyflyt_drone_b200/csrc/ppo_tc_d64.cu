// K4 for observations of 33..64 floats (the 56-float duck-only ObjLock observation, train/train_objlock.py:186-290): ppo_tc.cu
// compiled with a 64-wide layer-1 operand (two 32-column slabs per row).
#define PPO_A_BUILD 4
#define PPO_D_BUILD 64
#include "ppo_tc.cu"
