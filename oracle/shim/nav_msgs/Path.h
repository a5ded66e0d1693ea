#pragma once
#include "Odometry.h"
