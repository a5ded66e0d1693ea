#pragma once
#include "Header.h"
