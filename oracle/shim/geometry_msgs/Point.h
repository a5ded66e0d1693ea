#pragma once
#include "Pose.h"
