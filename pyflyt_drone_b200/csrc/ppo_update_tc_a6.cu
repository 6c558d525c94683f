// K6 for six-channel policies (the low-level env's MlpPolicy, train/train_lowlevel_cmd.py:97-110): ppo_update_tc.cu compiled
// with the action width set to 6.
#define PPO_A_BUILD 6
#include "ppo_update_tc.cu"
