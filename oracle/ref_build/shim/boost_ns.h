// TEST INFRASTRUCTURE ONLY: RasterDataBlock.cpp has `using namespace boost;`.
namespace boost {}
