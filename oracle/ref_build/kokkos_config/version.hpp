/* stands in for the src/version.hpp the reference's CMake writes (CMakeLists.txt:155-162) */
#define IDEFIX_GIT_COMMIT "oracle"
#define IDEFIX_VERSION_MAJOR "2"
#define IDEFIX_VERSION_MINOR "2"
#define IDEFIX_VERSION_PATCH "02"
#define IDEFIX_VERSION "2.2.02-oracle"
