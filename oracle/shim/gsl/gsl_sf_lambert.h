/* Oracle build stand-in: the reference includes this GSL header (aiolos.h:18) but every
 * GSL call is commented out (helpers.cpp:413-460). Test infrastructure only. */
