/* stand-in so the unmodified reference sources compile; everything lives in glm/glm.hpp */
#include <glm/glm.hpp>
