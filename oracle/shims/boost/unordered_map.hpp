// TEST INFRASTRUCTURE (oracle shim): Boost is absent; EulerUpstream.hpp:41 includes this
// header but uses nothing from it.
