#include "glm.hpp"
