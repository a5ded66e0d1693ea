#pragma once
#include "Image.h"
