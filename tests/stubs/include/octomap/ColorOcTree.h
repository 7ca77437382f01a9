#pragma once
#include <octomap/octomap.h>
