#pragma once
#include "Marker.h"
