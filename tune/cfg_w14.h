#define YF_CFGRES1_W WirbCfg<4,8,4,1,4,8,6,14,true,true>
#define YF_CFGRES2_W WirbCfg<8,32,8,1,8,2,6,8,true,false>
#define YF_CFGDOWN2_W WirbCfg<8,32,8,2,4,2,8,14,false,false>
