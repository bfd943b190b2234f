// placeholder until the keyframe store lands
#include "engine.cuh"
