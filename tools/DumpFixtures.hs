{-# LANGUAGE OverloadedStrings #-}
-- | Dumps reference-held value vectors for the QAP path as JSON, to be replayed by this repository's parity tests
-- (arithmetic-circuits_b200/json_io.py reads exactly these aeson encodings).  This is the one route to lifting the
-- "parity unpinned" caveat of DESIGN.md section 3: no Haskell toolchain exists in the build image, so the values below
-- have to be produced on a machine with GHC, from the UNMODIFIED reference:
--
-- >   cd arithmetic-circuits && cp /path/to/tools/DumpFixtures.hs . && stack runghc DumpFixtures.hs > reference_fixtures.json
--
-- then  python tools/replay_fixtures.py reference_fixtures.json  in this repository (GPU box) compares every value
-- bit for bit with what libacg.so computes.
--
-- Contents: for KAT-1 (test/Test/QAP.hs:48-62, Lagrange build on roots 7,8,9, and its FFT build) and KAT-3
-- (bench/Circuit.hs:17-24, FFT build with roots 0,1 and Example.hs roots 1,2): the circuit, the inputs, the assignment
-- of generateAssignment, the GenQAP, the QAP value, verificationWitness (h) and verificationWitnessZk 3 5 7.
module Main (main) where

import           Protolude

import           Circuit.Affine           (AffineCircuit (..))
import           Circuit.Arithmetic       (ArithCircuit (..), Gate (..), Wire (..))
import           Data.Aeson               (ToJSON, Value, encode, object, toJSON, (.=))
import qualified Data.ByteString.Lazy     as BL
import qualified Data.Map                 as Map
import           Data.Pairing.BN254       (Fr, getRootOfUnity)
import           QAP

kat1 :: ArithCircuit Fr
kat1 = ArithCircuit
  [ Mul (Var (InputWire 0)) (Var (InputWire 1)) (IntermediateWire 0)
  , Mul (Var (InputWire 2)) (Var (InputWire 3)) (IntermediateWire 1)
  , Mul (Add (ConstGate 10) (Var (IntermediateWire 0))) (Var (IntermediateWire 1)) (OutputWire 0)
  ]

kat3 :: ArithCircuit Fr
kat3 = ArithCircuit
  [ Mul (Var (InputWire 0)) (Var (InputWire 1)) (IntermediateWire 0)
  , Mul (Var (IntermediateWire 0)) (Add (Var (InputWire 0)) (Var (InputWire 2))) (OutputWire 0)
  ]

fixture :: Text -> ArithCircuit Fr -> Map Int Fr -> [[Fr]] -> Bool -> Value
fixture name circuit inputs roots lagrange = object
  [ "name"        .= name
  , "circuit"     .= circuit
  , "inputs"      .= inputs
  , "roots"       .= roots
  , "build"       .= (if lagrange then "arithCircuitToQAP" else "arithCircuitToQAPFFT getRootOfUnity" :: Text)
  , "assignment"  .= assignment
  , "gen_qap"     .= arithCircuitToGenQAP roots circuit
  , "qap"         .= qap
  , "verify"      .= verifyAssignment qap assignment
  , "h"           .= verificationWitness qap assignment
  , "h_zk_3_5_7"  .= verificationWitnessZk 3 5 7 qap assignment
  , "witness"     .= qapSetToMap assignment
  ]
  where
    assignment = generateAssignment circuit inputs
    qap | lagrange  = arithCircuitToQAP roots circuit
        | otherwise = arithCircuitToQAPFFT getRootOfUnity roots circuit

main :: IO ()
main = BL.putStr . encode $
  [ fixture "kat1_lagrange_roots_7_8_9" kat1 in1 [[7], [8], [9]] True
  , fixture "kat1_fft_roots_1_2_3"      kat1 in1 [[1], [2], [3]] False
  , fixture "kat3_bench_fft_roots_0_1"  kat3 in3 [[0], [1]] False
  , fixture "kat3_example_fft_roots_1_2" kat3 in3 [[1], [2]] False
  , fixture "kat3_lagrange_roots_0_1"   kat3 in3 [[0], [1]] True
  , object [ "name" .= ("roots_of_unity" :: Text), "getRootOfUnity" .= [ getRootOfUnity k :: Fr | k <- [0 .. 28] ] ]
  ]
  where
    in1 = Map.fromList [(0, 2), (1, 3), (2, 4), (3, 5)]
    in3 = Map.fromList [(0, 7), (1, 5), (2, 4)]
