{-# LANGUAGE ForeignFunctionInterface #-}
-- RaytracerB200.hs -- the reference-side binding for libblackstar_b200.so.
--
-- Drop-in for the two exports Main.doRender uses (app/Main.hs:109,116):
--     Raytracer.render   :: Config -> StarTree -> Image U RGB Double     (src/Raytracer.hs:53)
--     ImageFilters.bloom :: Double -> Int -> Image U RGB Double -> IO (Image U RGB Double)
-- NOT COMPILED in the repo that ships it: the build image has no ghc/stack/cabal.
-- Cabal change (blackstar.cabal library stanza, :16-42):
--     other-modules:   RaytracerB200
--     extra-libraries: blackstar_b200
--     build-depends:   ... , storable-record or hand-written Storable instances as below
module RaytracerB200 (withB200, renderB200, renderB200Word8, bloomB200, B200) where

import Foreign
import Foreign.C.Types
import Foreign.C.String
import qualified Data.Vector.Storable as VS
import Data.KdMap.Static (assocs)
import Data.Massiv.Array as A
import Graphics.ColorSpace
import Linear (V3(..))

import ConfigFile
import StarMap (StarTree)

data Ctx
newtype B200 = B200 (Ptr Ctx)

foreign import ccall safe "bsb_create"      c_create    :: CInt -> IO (Ptr Ctx)
foreign import ccall safe "bsb_destroy"     c_destroy   :: Ptr Ctx -> IO ()
foreign import ccall safe "bsb_last_error"  c_lastError :: Ptr Ctx -> IO CString
foreign import ccall safe "bsb_set_stars"   c_setStars  :: Ptr Ctx -> Ptr CStar -> CSize -> IO CInt
foreign import ccall safe "bsb_render_full" c_renderFull
    :: Ptr Ctx -> Ptr CCamera -> Ptr CScene -> Ptr CFloat -> Ptr () -> IO CInt
foreign import ccall safe "bsb_render_full_srgb8" c_renderFullSrgb8
  :: Ptr Ctx -> Ptr CCamera -> Ptr CScene -> Ptr Word8 -> Ptr () -> IO CInt
foreign import ccall safe "bsb_render"      c_render
    :: Ptr Ctx -> Ptr CCamera -> Ptr CScene -> CInt -> CInt -> Ptr CFloat -> Ptr () -> IO CInt
foreign import ccall safe "bsb_bloom"       c_bloom
    :: Ptr Ctx -> CDouble -> CInt -> CInt -> CInt -> Ptr CFloat -> Ptr CFloat -> IO CInt

-- bsb_star: { double pos[3]; double hue; double sat; int32 mag; int32 pad } = 48 bytes
data CStar = CStar !(V3 Double) !Double !Double !Int32
instance Storable CStar where
  sizeOf _ = 48; alignment _ = 8
  peek _ = error "CStar: write-only"
  poke p (CStar (V3 x y z) h s m) = do
    pokeByteOff p 0 x; pokeByteOff p 8 y; pokeByteOff p 16 z
    pokeByteOff p 24 h; pokeByteOff p 32 s; pokeByteOff p 40 m; pokeByteOff p 44 (0 :: Int32)

-- bsb_camera: pos[3], look_at[3], up[3], fov = 80 bytes
newtype CCamera = CCamera Camera
instance Storable CCamera where
  sizeOf _ = 80; alignment _ = 8
  peek _ = error "CCamera: write-only"
  poke p (CCamera c) = do
    let v3 o (V3 x y z) = pokeByteOff p o x >> pokeByteOff p (o+8) y >> pokeByteOff p (o+16) z
    v3 0 (position c); v3 24 (lookAt c); v3 48 (upVec c); pokeByteOff p 72 (fov c)

-- bsb_scene: 10 doubles then 4 int32 = 96 bytes (field order of include/blackstar_b200.h)
newtype CScene = CScene Scene
instance Storable CScene where
  sizeOf _ = 96; alignment _ = 8
  peek _ = error "CScene: write-only"
  poke p (CScene s) = do
    let PixelHSI h sa i = diskColor s          -- hue already / 360 (src/ConfigFile.hs:51)
        (w, hgt) = resolution s
    pokeByteOff p 0  (stepSize s);      pokeByteOff p 8  (bloomStrength s)
    pokeByteOff p 16 (starIntensity s); pokeByteOff p 24 (starSaturation s)
    pokeByteOff p 32 h; pokeByteOff p 40 sa; pokeByteOff p 48 i
    pokeByteOff p 56 (diskOpacity s);   pokeByteOff p 64 (diskInner s); pokeByteOff p 72 (diskOuter s)
    pokeByteOff p 80 (fromIntegral (bloomDivider s) :: Int32)
    pokeByteOff p 84 (fromIntegral w :: Int32); pokeByteOff p 88 (fromIntegral hgt :: Int32)
    pokeByteOff p 92 (if supersampling s then 1 else 0 :: Int32)

check :: Ptr Ctx -> CInt -> IO ()
check _ 0 = return ()
check ctx rc = do msg <- peekCString =<< c_lastError ctx
                  ioError (userError ("blackstar_b200 [" ++ show rc ++ "]: " ++ msg))

-- | Create a context on every visible GPU, upload the star map once (replaces the StarTree
--   argument: the k-d tree is rebuilt on the device side from the flat association list).
withB200 :: StarTree -> (B200 -> IO a) -> IO a
withB200 tree act = do
  ctx <- c_create 0
  if ctx == nullPtr
    then do msg <- peekCString =<< c_lastError nullPtr; ioError (userError msg)
    else do
      let stars = VS.fromList [ CStar p h s (fromIntegral m) | (p, (m, h, s)) <- assocs tree ]
      VS.unsafeWith stars $ \sp -> check ctx =<< c_setStars ctx sp (fromIntegral (VS.length stars))
      r <- act (B200 ctx)
      c_destroy ctx
      return r

floatsToImage :: Int -> Int -> VS.Vector CFloat -> Image U RGB Double
floatsToImage w h v = makeArrayR U Par (h :. w) $ \(y :. x) ->
  let o = 4 * (y * w + x); f k = realToFrac (v VS.! (o + k)) in PixelRGB (f 0) (f 1) (f 2)

-- | Raytracer.render (incl. supersample) + ImageFilters.bloom when bloomStrength /= 0,
--   i.e. everything Main.doRender does before writeImg (app/Main.hs:105-118).
renderB200 :: B200 -> Config -> IO (Image U RGB Double)
renderB200 (B200 ctx) cfg = do
  let (w, h) = resolution (scene cfg)
  buf <- mallocForeignPtrArray (4 * w * h) :: IO (ForeignPtr CFloat)
  with (CCamera (camera cfg)) $ \cp -> with (CScene (scene cfg)) $ \sp ->
    withForeignPtr buf $ \bp -> check ctx =<< c_renderFull ctx cp sp bp nullPtr
  return $ floatsToImage w h (VS.unsafeFromForeignPtr0 buf (4 * w * h))

-- | Everything Main.doRender does INCLUDING writeImg's pixel map (app/Main.hs:105-123,
--   src/Raytracer.hs:23-32): render, supersample, bloom, sRGB, toWord8 -- the RGB8 image the PNG encoder
--   takes.  The sRGB + toWord8 map runs in the epilogue of the last bloom launch, the float frame is never
--   written and only 3 bytes per pixel cross PCIe; `doRender` then becomes
--       img8 <- timeAction "Rendering" =<< renderB200Word8 b200 cfg
--       writeArray PNG def outName img8        -- instead of writeImg (no `A.map (toWord8 . fmap sRGB)`)
--   This is the call bench.py times as `e2e` (with the same plain-malloc buffer mallocForeignPtrArray gives).
renderB200Word8 :: B200 -> Config -> IO (Image S RGB Word8)
renderB200Word8 (B200 ctx) cfg = do
  let (w, h) = resolution (scene cfg)
  buf <- mallocForeignPtrArray (3 * w * h) :: IO (ForeignPtr Word8)
  with (CCamera (camera cfg)) $ \cp -> with (CScene (scene cfg)) $ \sp ->
    withForeignPtr buf $ \bp -> check ctx =<< c_renderFullSrgb8 ctx cp sp bp nullPtr
  let v = VS.unsafeFromForeignPtr0 buf (3 * w * h)
  return $ makeArrayR S Seq (h :. w) $ \(y :. x) ->
    let o = 3 * (y * w + x) in PixelRGB (v VS.! o) (v VS.! (o + 1)) (v VS.! (o + 2))

-- | ImageFilters.bloom strength divider img
bloomB200 :: B200 -> Double -> Int -> Image U RGB Double -> IO (Image U RGB Double)
bloomB200 (B200 ctx) strength divider img = do
  let (h :. w) = size img
      flat = VS.fromList (concat [ [realToFrac r, realToFrac g, realToFrac b, 1]
                                 | PixelRGB r g b <- A.toList img ]) :: VS.Vector CFloat
  out <- mallocForeignPtrArray (4 * w * h)
  VS.unsafeWith flat $ \ip -> withForeignPtr out $ \op ->
    check ctx =<< c_bloom ctx (realToFrac strength) (fromIntegral divider) (fromIntegral w) (fromIntegral h) ip op
  return $ floatsToImage w h (VS.unsafeFromForeignPtr0 out (4 * w * h))
