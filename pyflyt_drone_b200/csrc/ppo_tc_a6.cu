// K4 for six-channel policies (the low-level env's MlpPolicy): ppo_tc.cu compiled with the action width set to 6.
#define PPO_A_BUILD 6
#include "ppo_tc.cu"
